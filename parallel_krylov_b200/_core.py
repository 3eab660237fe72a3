"""Host side of the drop-in boundary: buffers are torch CUDA tensors, compute is libpkrylov (ctypes).

Mirrors the helper layer of the reference (``init`` / ``MultiGpu`` in /root/reference/v3/gpu/common.py:25-126):
``Context`` ≙ ``MultiGpu.init`` (one device, one stream), ``Operator`` ≙ ``MultiGpu.alloc`` (A uploaded once, kernel
chosen from its nnz distribution), ``solve`` ≙ the body of the five v3 solver functions.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import PkError, SolveOpts, SolveResult, check

# History entries kept on the device when maxiter is larger (the reference allocates maxiter+1 doubles + ints up front,
# /root/reference/v3/cpu/common.py:33-34 — 3 GiB for maxiter = N at 512^3).  Entries beyond the cap are dropped, never
# written; the solve itself is unaffected, info["history_truncated"] says so, PK_HIST_CAP overrides the cap.
HIST_CAP = int(os.environ.get("PK_HIST_CAP", 1 << 22))


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _require_cuda():
    if not torch.cuda.is_available():
        raise PkError("parallel_krylov_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


class Context:
    """One per device: stream, reduction scratch, device-resident solver state, optional NCCL communicator."""

    _by_device: dict = {}

    def __init__(self, device: int):
        _require_cuda()
        self.lib = _lib.load()
        self.device = int(device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.pk_ctx_create(C.byref(h), self.device, C.c_void_p(0)), "pk_ctx_create")
        self.handle = h
        self.n_ranks = 1
        self.rank = 0
        self.group = None

    @classmethod
    def get(cls, device: Optional[int] = None) -> "Context":
        _require_cuda()
        if device is None:
            device = torch.cuda.current_device()
        device = int(device)
        if device not in cls._by_device:
            cls._by_device[device] = Context(device)
        return cls._by_device[device]

    @property
    def torch_device(self):
        return torch.device("cuda", self.device)

    def sync(self):
        check(self.lib.pk_ctx_sync(self.handle), "pk_ctx_sync")

    # -- communicator ---------------------------------------------------------------------------------------
    def init_comm(self, group=None):
        """Create the NCCL communicator libpkrylov uses, bootstrapped over an existing torch.distributed group
        (≙ ``MultiGpu.joint_mpi(comm)``, /root/reference/v3/gpu/mpi/common.py:168-171)."""
        import torch.distributed as dist
        if self.n_ranks > 1:
            return
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        if world == 1:
            return
        path = _find_nccl().encode()
        idbuf = C.create_string_buffer(_lib.PK_NCCL_ID_BYTES)
        if rank == 0:
            check(self.lib.pk_nccl_unique_id(path, idbuf), "pk_nccl_unique_id")
        backend = dist.get_backend(group)
        dev = self.torch_device if backend == "nccl" else torch.device("cpu")
        t = torch.tensor(list(idbuf.raw), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(t.cpu().tolist())
        with torch.cuda.device(self.device):
            check(self.lib.pk_comm_init(self.handle, path, world, rank, raw), "pk_comm_init")
        self.n_ranks, self.rank, self.group = world, rank, group
        self.fused_allreduce = False
        if os.environ.get("PK_ALLREDUCE", "p2p") != "nccl":
            self._open_mailboxes(group, world, rank, dev)

    def _open_mailboxes(self, group, world, rank, dev):
        """Exchange CUDA-IPC handles of the per-rank mailboxes so that the reducing kernels can all-reduce their dot
        products themselves over NVLink (csrc/pk_device.cuh).  Falls back to ncclAllReduce if mapping fails."""
        import torch.distributed as dist
        hb = C.create_string_buffer(_lib.PK_IPC_HANDLE_BYTES)
        with torch.cuda.device(self.device):
            check(self.lib.pk_p2p_handle(self.handle, hb), "pk_p2p_handle")
        mine = torch.tensor(list(hb.raw), dtype=torch.uint8, device=dev)
        allh = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine, group=group)
        raw = b"".join(bytes(t.cpu().tolist()) for t in allh)
        with torch.cuda.device(self.device):
            rc = self.lib.pk_p2p_open(self.handle, world, rank, raw)
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 1:
            self.fused_allreduce = True
        else:   # all ranks must agree on the path
            raise PkError("peer mapping of the all-reduce mailboxes failed on some rank; set PK_ALLREDUCE=nccl")


def _find_nccl() -> str:
    """The libnccl torch already mapped into this process (so both share one NCCL), else the wheel's copy."""
    try:
        with open("/proc/self/maps") as fh:
            for line in fh:
                if "libnccl" in line:
                    return line.split()[-1]
    except OSError:
        pass
    try:
        import nvidia.nccl as nn
        cand = os.path.join(os.path.dirname(nn.__file__), "lib", "libnccl.so.2")
        if os.path.exists(cand):
            return cand
    except Exception:
        pass
    return "libnccl.so.2"


class Operator:
    """A (row block of) A resident in HBM: CSR (int32 indices, fp64 values) or dense row-major fp64."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.handle = C.c_void_p()
        self.n_rows = 0          # local rows
        self.n_global = 0
        self.row0 = 0
        self.tensors = {}        # keeps the device arrays alive
        self.kind = None
        self.nnz = 0
        self.h2d_bytes = 0
        self.n_halo = 0
        self.row_offsets = None
        self.compressed = False
        self.n_patterns = 0

    def __del__(self):
        try:
            if self.handle:
                self.ctx.lib.pk_mat_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    @property
    def ld(self) -> int:
        return int(self.ctx.lib.pk_mat_ld(self.handle))

    def kernel_info(self):
        k, r, c = C.c_int(), C.c_int(), C.c_int()
        check(self.ctx.lib.pk_mat_kernel_info(self.handle, C.byref(k), C.byref(r), C.byref(c)))
        return {"kernel": "row-pattern" if self.compressed else ["csr-stream", "csr-vector", "dense-gemv"][k.value],
                "tile_rows": r.value, "tile_cap": c.value, "patterns": self.n_patterns}

    def matpow_info(self, k: int):
        """Which one-pass basis kernel a k-skip solve with this k would use (csrc/pk_matpow.cu)."""
        kind, win = C.c_int(), C.c_int()
        check(self.ctx.lib.pk_mat_matpow_info(self.ctx.handle, self.handle, int(k), C.byref(kind), C.byref(win)))
        return {"kernel": ["none", "general", "dense-band"][kind.value], "window_rows": win.value}

    # -- constructors -------------------------------------------------------------------------------------------
    @classmethod
    def from_csr_tensors(cls, rowptr: torch.Tensor, col: torch.Tensor, val: torch.Tensor, n_cols: int,
                         ctx: Optional[Context] = None) -> "Operator":
        ctx = ctx or Context.get(rowptr.device.index)
        dev = ctx.torch_device
        op = cls(ctx)
        h2d = 0
        def dv(t, dtype):
            nonlocal h2d
            if t.device != dev:
                h2d += t.numel() * torch.empty((), dtype=dtype).element_size()
            if dtype == torch.int32 and t.dtype != torch.int32 and t.numel():
                # a wider index type must FIT before it is narrowed: a silent wrap would pass the device-side range check
                lo, hi = int(t.min()), int(t.max())
                if lo < 0 or hi >= 2 ** 31:
                    raise PkError(f"CSR index out of the int32 range ({lo}..{hi}); a block must have nnz < 2^31 and "
                                  "fewer than 2^31 columns")
            return t.to(device=dev, dtype=dtype).contiguous()
        if int(n_cols) >= 2 ** 31:
            raise PkError("a block must have fewer than 2^31 columns (int32 column indices)")
        # 64-bit row pointers (nnz >= 2^31): kept as int64, the library applies the block in row segments of < 2^31
        # nonzeros (PK_SEG_NNZ forces that path with small segments: tests)
        wide = rowptr.dtype == torch.int64 and rowptr.numel() > 0 and (
            int(rowptr[-1]) >= 2 ** 31 or bool(os.environ.get("PK_SEG_NNZ")))
        rowptr = dv(rowptr, torch.int64 if wide else torch.int32)
        col, val = dv(col, torch.int32), dv(val, torch.float64)
        op.tensors = {"rowptr": rowptr, "col": col, "val": val}
        op.n_rows = rowptr.numel() - 1
        op.n_global = int(n_cols)
        op.nnz = int(val.numel())
        op.h2d_bytes = h2d
        op.kind = "csr"
        op.index64 = wide
        with torch.cuda.device(ctx.device):
            fn = ctx.lib.pk_mat_csr64 if wide else ctx.lib.pk_mat_csr
            check(fn(ctx.handle, C.byref(op.handle), op.n_rows, int(n_cols), op.nnz, _ptr(rowptr), _ptr(col), _ptr(val)),
                  "pk_mat_csr64" if wide else "pk_mat_csr")
        return op

    @classmethod
    def from_dense_tensor(cls, a: torch.Tensor, ctx: Optional[Context] = None) -> "Operator":
        ctx = ctx or Context.get(a.device.index if a.is_cuda else None)
        op = cls(ctx)
        op.h2d_bytes = 0 if a.is_cuda else a.numel() * 8
        a = a.to(device=ctx.torch_device, dtype=torch.float64).contiguous()
        op.tensors = {"dense": a}
        op.n_rows, op.n_global = int(a.shape[0]), int(a.shape[1])
        op.nnz = a.numel()
        op.kind = "dense"
        with torch.cuda.device(ctx.device):
            check(ctx.lib.pk_mat_dense(ctx.handle, C.byref(op.handle), op.n_rows, op.n_global, _ptr(a),
                                       int(a.stride(0))), "pk_mat_dense")
        return op

    @classmethod
    def from_any(cls, A, ctx: Optional[Context] = None) -> "Operator":
        """scipy sparse (any format) / numpy 2-D / torch dense or sparse-CSR / (rowptr, col, val, n) / Operator —
        the inputs ``MultiGpu.alloc`` accepts (/root/reference/v3/gpu/common.py:100-104) plus device-resident ones."""
        if isinstance(A, Operator):
            return A
        ctx = ctx or Context.get()
        if isinstance(A, np.ndarray):
            if A.ndim != 2:
                raise PkError("dense A must be 2-D")
            return cls.from_dense_tensor(torch.from_numpy(np.ascontiguousarray(A, dtype=np.float64)), ctx)
        if isinstance(A, torch.Tensor):
            if A.layout == torch.sparse_csr:
                return cls.from_csr_tensors(A.crow_indices(), A.col_indices(), A.values(), A.shape[1], ctx)
            return cls.from_dense_tensor(A, ctx)
        if isinstance(A, (tuple, list)) and len(A) == 4:
            rp, ci, va, n = A
            as_t = lambda v: v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
            return cls.from_csr_tensors(as_t(rp), as_t(ci), as_t(va), int(n), ctx)
        if hasattr(A, "tocsr"):
            m = A.tocsr()
            if not m.has_sorted_indices:
                m = m.sorted_indices()
            return cls.from_csr_tensors(torch.from_numpy(np.ascontiguousarray(m.indptr)),
                                        torch.from_numpy(np.ascontiguousarray(m.indices)),
                                        torch.from_numpy(np.ascontiguousarray(m.data, dtype=np.float64)),
                                        m.shape[1], ctx)
        raise PkError(f"unsupported matrix type {type(A)!r}")

    # -- opt-in lossless compression ------------------------------------------------------------------------------
    def compress_patterns(self) -> bool:
        """Group rows by their (column offset from the row, value) sequence and, if the table of distinct patterns is
        small (≤ 32767 patterns, ≤ 64 KiB), switch the operator to the pattern kernel: one 16-bit id per row instead of
        12 bytes per nonzero, same products in the same order (bit-identical results).  Returns whether it applied.
        Constant-coefficient stencils compress (27 patterns for the 3-D 7-point Laplacian); matrices with distinct
        values per row (e.g. ``problems.banded_spd``) do not and keep the CSR kernels.  The device verifies every row
        against its pattern before the switch."""
        if self.kind != "csr" or self.n_rows == 0 or self.compressed or getattr(self, "index64", False):
            return self.compressed
        ctx, dev, n = self.ctx, self.ctx.torch_device, self.n_rows
        rowptr, col, val = self.tensors["rowptr"], self.tensors["col"], self.tensors["val"]
        hashes = torch.empty(n, dtype=torch.int64, device=dev)
        torch.cuda.current_stream(ctx.device).synchronize()
        with torch.cuda.device(ctx.device):
            check(ctx.lib.pk_mat_row_hashes(self.handle, _ptr(hashes)), "pk_mat_row_hashes")
            ctx.sync()
            uniq, inv = torch.unique(hashes, return_inverse=True)
            del hashes
            n_pat = int(uniq.numel())
            if n_pat > 32767:
                return False
            rep = torch.full((n_pat,), n, dtype=torch.int64, device=dev)
            rep.scatter_reduce_(0, inv, torch.arange(n, device=dev), reduce="amin")      # first row of each pattern
            rp64 = rowptr.to(torch.int64)
            lens = rp64[rep + 1] - rp64[rep]
            n_ent = int(lens.sum().item())
            if 12 * n_ent + 4 * (n_pat + 1) > 64 * 1024:
                return False
            ptr = torch.zeros(n_pat + 1, dtype=torch.int64, device=dev)
            torch.cumsum(lens, 0, out=ptr[1:])
            seg = torch.repeat_interleave(torch.arange(n_pat, device=dev), lens)
            src = rp64[rep[seg]] + (torch.arange(n_ent, device=dev) - ptr[seg])
            off = (col[src].to(torch.int64) - rep[seg]).to(torch.int32).contiguous()
            pval = val[src].contiguous()
            ids = inv.to(torch.int16).contiguous()               # < 32768 patterns: the bits are the uint16 id
            del inv
            ptr32 = ptr.to(torch.int32).contiguous()
            rc = ctx.lib.pk_mat_set_patterns(self.handle, n_pat, n_ent, _ptr(ids), _ptr(ptr32), _ptr(off), _ptr(pval))
        if rc != 0:
            return False
        self.tensors.update({"pat_id": ids, "pat_ptr": ptr32, "pat_off": off, "pat_val": pval})
        self.compressed = True
        self.n_patterns = n_pat
        return True

    # -- building blocks (tests / benchmarks) ------------------------------------------------------------------
    def matvec(self, x: torch.Tensor, dot_with: Optional[torch.Tensor] = None, x1: Optional[torch.Tensor] = None):
        """y = A x through the solver's operator kernel.  Returns y (and y1), plus the fused sums when asked."""
        ctx = self.ctx
        ld = self.ld
        def padded(v):
            buf = torch.zeros(ld, dtype=torch.float64, device=ctx.torch_device)
            buf[: v.numel()] = v.to(ctx.torch_device, torch.float64)
            return buf
        xb = padded(x)
        x1b = padded(x1) if x1 is not None else None
        y = torch.empty(ld, dtype=torch.float64, device=ctx.torch_device)
        y1 = torch.empty(ld, dtype=torch.float64, device=ctx.torch_device) if x1 is not None else None
        w = dot_with.to(ctx.torch_device, torch.float64).contiguous() if dot_with is not None else None
        sums = torch.zeros(3, dtype=torch.float64, device=ctx.torch_device)
        torch.cuda.current_stream(ctx.device).synchronize()
        with torch.cuda.device(ctx.device):
            check(ctx.lib.pk_spmv(ctx.handle, self.handle, _ptr(xb), _ptr(y), _ptr(x1b), _ptr(y1), _ptr(w),
                                  _ptr(sums)), "pk_spmv")
        ctx.sync()
        out = [y[: self.n_rows]]
        if x1 is not None:
            out.append(y1[: self.n_rows])
        if dot_with is not None:
            out.append(sums)
        return out[0] if len(out) == 1 else tuple(out)


# ------------------------------------------------------------------------------------------------------------------
def _banner_start(name: str, k):
    # same fields as the reference banner (/root/reference/v3/common.py:2-6)
    print("# " + "=" * 16 + " INFO " + "=" * 16 + " #")
    print(f"Method:\t\t{name}")
    if k is not None:
        print(f"Initial_k:\t{k}")


def _banner_finish(elapsed, converged, iters, final_res, final_k=None):
    # /root/reference/v3/common.py:9-23
    print(f"Time:\t\t{elapsed} s")
    print(f"Status:\t\t{'converged' if converged else 'diverged'}")
    print(f"Iteration:\t{iters} times")
    print(f"Final_Residual:\t{final_res}")
    if final_k:
        print(f"Final_k:\t{final_k}")
    print("# " + "=" * 38 + " #")


_NAMES = {"cg": "CG", "mrr": "MrR", "kskipcg": "k-skip CG", "kskipmrr": "k-skip MrR",
          "adaptivekskipmrr": "Adaptive k-skip MrR", "cgcg": "chronopoulos gear"}


def quiet() -> bool:
    return os.environ.get("PK_QUIET", "0") not in ("0", "")


def solve(method: str, A, b, x=None, tol=1e-05, maxiter=None, k=0, *, check_every: int = 0,
          use_graph: Optional[bool] = None, verbose: Optional[bool] = None, ctx: Optional[Context] = None,
          compress: Optional[bool] = None, M=None, basis=None):
    """Shared body of the five entry points.  Returns ``(x, info)`` like the reference
    (/root/reference/v3/gpu/cg.py:47-52): x and the histories are torch CUDA tensors."""
    op = Operator.from_any(A, ctx)
    ctx = op.ctx
    dev = ctx.torch_device
    lib = ctx.lib
    n = op.n_rows
    if compress is None:
        compress = os.environ.get("PK_COMPRESS", "0") not in ("0", "")
    if compress:
        op.compress_patterns()           # opt-in, lossless; silently keeps CSR when the matrix has no repeating rows
    if op.n_global != n and op.row_offsets is None:
        raise PkError(f"A must be square (got {n} x {op.n_global}); use parallel_krylov_b200.mpi for row blocks")
    ld = op.ld

    # init(): /root/reference/v3/gpu/common.py:25-40
    h2d = op.h2d_bytes
    if isinstance(b, np.ndarray):
        h2d += b.size * 8
        b_t = torch.from_numpy(np.ascontiguousarray(b, dtype=np.float64)).to(dev)
    else:
        if not b.is_cuda:
            h2d += b.numel() * 8
        b_t = b.to(device=dev, dtype=torch.float64).contiguous()
    if b_t.numel() != n:
        raise PkError(f"b has {b_t.numel()} entries, A has {n} rows")
    x_t = torch.zeros(ld, dtype=torch.float64, device=dev)
    x_is_zero = True
    if isinstance(x, np.ndarray):           # only an ndarray counts as an initial guess (v3/gpu/common.py:30-33)
        x_t[:n] = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
        h2d += x.size * 8
        x_is_zero = False
    elif isinstance(x, torch.Tensor):
        x_t[:n] = x.to(device=dev, dtype=torch.float64)
        x_is_zero = False
    if maxiter is None:
        maxiter = op.n_global               # v3/gpu/common.py:35-36
    maxiter = int(maxiter)
    k = int(k)
    if not 0 <= k <= _lib.PK_KMAX:
        raise PkError(f"k must be in 0..{_lib.PK_KMAX}")
    mid = _lib.METHOD_IDS[method]
    hist_len = max(4, min(maxiter + k + 3, HIST_CAP))
    residual = torch.zeros(hist_len, dtype=torch.float64, device=dev)
    nosl = torch.zeros(hist_len, dtype=torch.int64, device=dev)
    khist = torch.zeros(hist_len, dtype=torch.int64, device=dev) if method == "adaptivekskipmrr" else None
    nwork = int(lib.pk_work_doubles(mid, ld, k))
    work = torch.zeros(nwork, dtype=torch.float64, device=dev)

    # Jacobi preconditioner of the Chronopoulos-Gear entry point: M = "jacobi" (diag(A), extracted on the device) or the
    # diagonal itself (local rows).  The five v3 methods accept and ignore M like the reference (v3/cpu/cg.py:7).
    mdiag = None
    if method == "cgcg" and M is not None:
        if isinstance(M, str):
            if M != "jacobi":
                raise PkError(f"unknown preconditioner {M!r} (None, 'jacobi' or the diagonal of M)")
            mdiag = torch.empty(n, dtype=torch.float64, device=dev)
            with torch.cuda.device(ctx.device):
                check(lib.pk_mat_diagonal(op.handle, _ptr(mdiag)), "pk_mat_diagonal")
        else:
            mt = torch.from_numpy(np.ascontiguousarray(M, dtype=np.float64)) if isinstance(M, np.ndarray) else M
            if mt.numel() == op.n_global and op.n_global != n:
                mt = mt[op.row0:op.row0 + n]
            if mt.numel() != n:
                raise PkError(f"M has {mt.numel()} entries; expected the diagonal for {n} local rows")
            mdiag = mt.to(device=dev, dtype=torch.float64).contiguous()
    # Basis of the k-skip trips (kskipmrr, kskipcg): None / "monomial" = the reference's A^j r (parity path, default);
    # "chebyshev" = T_j((A - d)/c) r on Gershgorin bounds of the spectrum, or ("chebyshev", lam_lo, lam_hi) with bounds
    # of the caller's — numerically safe at k = 8, 12, 16 where the monomial basis is not (opt-in).
    basis_id, lam_lo, lam_hi = 0, 0.0, 0.0
    if basis is not None and basis != "monomial":
        name = basis if isinstance(basis, str) else basis[0]
        if name != "chebyshev" or method not in ("kskipmrr", "kskipcg"):
            raise PkError(f"basis={basis!r}: only 'chebyshev', for kskipmrr / kskipcg (or None / 'monomial')")
        basis_id = 1
        if isinstance(basis, str):
            bounds = (C.c_double * 2)()
            with torch.cuda.device(ctx.device):
                check(lib.pk_mat_gershgorin(op.handle, bounds), "pk_mat_gershgorin")
            lam_lo, lam_hi = float(bounds[0]), float(bounds[1])
            if ctx.n_ranks > 1:
                import torch.distributed as dist
                t = torch.tensor([-lam_lo, lam_hi], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=ctx.group)
                lam_lo, lam_hi = -float(t[0].item()), float(t[1].item())
        else:
            lam_lo, lam_hi = float(basis[1]), float(basis[2])
        if not lam_hi > lam_lo:
            raise PkError(f"Chebyshev basis: need lam_lo < lam_hi, got [{lam_lo}, {lam_hi}]")
    opts = SolveOpts(maxiter=maxiter, tol=float(tol), k=k, check_every=int(check_every),
                     use_graph=1 if (use_graph if use_graph is not None else True) else 0,
                     x_is_zero=1 if x_is_zero else 0, global_n=op.n_global,
                     d_mdiag=mdiag.data_ptr() if mdiag is not None else None,
                     basis=basis_id, lam_lo=lam_lo, lam_hi=lam_hi)
    res = SolveResult()
    show = (not quiet()) if verbose is None else verbose
    if show and ctx.rank == 0:
        _banner_start(_NAMES[method], k if method.startswith(("kskip", "adaptive")) else None)
    torch.cuda.current_stream(ctx.device).synchronize()
    t0 = time.perf_counter()
    with torch.cuda.device(ctx.device):
        check(lib.pk_solve(ctx.handle, mid, op.handle, _ptr(b_t), _ptr(x_t), _ptr(work), _ptr(residual), _ptr(nosl),
                           _ptr(khist), hist_len, C.byref(opts), C.byref(res)), f"pk_solve({method})")
    wall = time.perf_counter() - t0
    entries = int(min(res.entries, hist_len))
    if show and ctx.rank == 0:
        _banner_finish(res.elapsed_s, bool(res.converged), int(res.iterations), res.final_residual,
                       res.final_k if method == "adaptivekskipmrr" else None)
    info = {
        "time": res.elapsed_s,                 # loop time (CUDA events), the reference's timer placement
        "nosl": nosl[:entries],
        "residual": residual[:entries],
        # extras (not in the reference's dict)
        "converged": bool(res.converged),
        "iterations": int(res.iterations),
        "wall_time": wall,
        "gpu_launches": int(res.kernel_launches),
        "spmv": int(res.spmv_count),
        "h2d_bytes": int(h2d),
        "history_truncated": bool(res.entries > hist_len),
    }
    if info["history_truncated"] and ctx.rank == 0:
        print(f"parallel_krylov_b200: residual history truncated to {hist_len} of {int(res.entries)} entries "
              "(PK_HIST_CAP raises the cap)", file=sys.stderr)
    if method == "adaptivekskipmrr":
        info["khistory"] = khist[:entries]
        info["final_k"] = int(res.final_k)
    return x_t[:n], info

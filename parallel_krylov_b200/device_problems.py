"""Synthetic systems generated directly in HBM (no host copy): same matrices as ``problems.py``, bit for bit.
Used by bench.py at the BASELINE.json sizes (256³, 512³, N = 2^25) where a host build + upload would dominate."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from ._core import Context, _ptr
from ._lib import check


def _scan(counts: torch.Tensor) -> torch.Tensor:
    rowptr = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=counts.device)
    torch.cumsum(counts, 0, dtype=torch.int32, out=rowptr[1:])
    return rowptr


def stencil_csr(nx: int, ny: int, nz: int = 1, row0: int = 0, n_rows: Optional[int] = None,
                ctx: Optional[Context] = None):
    """(rowptr int32, col int32 GLOBAL, val f64, n_global) of rows [row0, row0+n_rows) of the 5-/7-point Laplacian."""
    ctx = ctx or Context.get()
    n = nx * ny * nz
    n_rows = n - row0 if n_rows is None else n_rows
    dev = ctx.torch_device
    counts = torch.empty(n_rows, dtype=torch.int32, device=dev)
    torch.cuda.current_stream(ctx.device).synchronize()
    with torch.cuda.device(ctx.device):
        check(ctx.lib.pk_gen_stencil_counts(ctx.handle, nx, ny, nz, row0, n_rows, _ptr(counts)))
        ctx.sync()
        rowptr = _scan(counts)
        del counts
        nnz = int(rowptr[-1].item())
        col = torch.empty(nnz, dtype=torch.int32, device=dev)
        val = torch.empty(nnz, dtype=torch.float64, device=dev)
        torch.cuda.current_stream(ctx.device).synchronize()
        check(ctx.lib.pk_gen_stencil_fill(ctx.handle, nx, ny, nz, row0, n_rows, _ptr(rowptr), _ptr(col), _ptr(val)))
        ctx.sync()
    return rowptr, col, val, n


def banded_csr(n: int, half_bw: int = 13, seed: int = 0, row0: int = 0, n_rows: Optional[int] = None,
               ctx: Optional[Context] = None):
    ctx = ctx or Context.get()
    n_rows = n - row0 if n_rows is None else n_rows
    dev = ctx.torch_device
    counts = torch.empty(n_rows, dtype=torch.int32, device=dev)
    torch.cuda.current_stream(ctx.device).synchronize()
    with torch.cuda.device(ctx.device):
        check(ctx.lib.pk_gen_banded_counts(ctx.handle, n, half_bw, row0, n_rows, _ptr(counts)))
        ctx.sync()
        rowptr = _scan(counts)
        del counts
        nnz = int(rowptr[-1].item())
        col = torch.empty(nnz, dtype=torch.int32, device=dev)
        val = torch.empty(nnz, dtype=torch.float64, device=dev)
        torch.cuda.current_stream(ctx.device).synchronize()
        check(ctx.lib.pk_gen_banded_fill(ctx.handle, n, half_bw, C.c_uint64(seed), row0, n_rows, _ptr(rowptr),
                                         _ptr(col), _ptr(val)))
        ctx.sync()
    return rowptr, col, val, n


def hash_normal(seed: int, n: int, offset: int = 0, ctx: Optional[Context] = None) -> torch.Tensor:
    ctx = ctx or Context.get()
    out = torch.empty(n, dtype=torch.float64, device=ctx.torch_device)
    torch.cuda.current_stream(ctx.device).synchronize()
    with torch.cuda.device(ctx.device):
        check(ctx.lib.pk_fill_hash_normal(ctx.handle, C.c_uint64(seed), offset, n, _ptr(out)))
        ctx.sync()
    return out

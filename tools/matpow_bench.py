"""Time the one-pass matrix-powers kernel and a k-skip MrR solve on the banded system of BASELINE.json configs[3]
(n = 2^25, 27 diagonals, k = 8) (kernel: host clock around synchronised groups of launches; solve: CUDA events).  The kernel variant is chosen by environment variables read once per
process (PK_MATPOW_BAND, PK_MATPOW), so run one process per variant:

    PK_MATPOW_BAND=0 python tools/matpow_bench.py      # general kernel (one row per thread)
    python tools/matpow_bench.py                       # dense-band kernel (two rows per thread)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parallel_krylov_b200 import _lib, device_problems as dp          # noqa: E402
from parallel_krylov_b200._core import Context, Operator, _ptr, solve  # noqa: E402


def main():
    n = int(os.environ.get("MP_N", 1 << 25))
    bw, k = int(os.environ.get("MP_BW", 13)), int(os.environ.get("MP_K", 8))
    ctx = Context.get(0)
    rp, col, val, _ = dp.banded_csr(n, bw, 0, row0=0, n_rows=n, ctx=ctx)
    op = Operator.from_csr_tensors(rp, col, val, n, ctx)
    b = dp.hash_normal(0, n, offset=0, ctx=ctx)
    ld = op.ld
    out = {"n": n, "bw": bw, "k": k, "env": {e: os.environ.get(e) for e in ("PK_MATPOW", "PK_MATPOW_BAND")}}
    if os.environ.get("PK_MATPOW", "1") not in ("0", ""):
        out["kernel"] = op.matpow_info(k)
        U = torch.zeros((k + 1) * ld, dtype=torch.float64, device="cuda")
        V = torch.zeros((k + 1) * ld, dtype=torch.float64, device="cuda")
        U[:n] = b
        V[:n] = b * 0.5
        # millisecond-scale launches on the library's own stream: host clock around a synchronised group of 5
        import time
        for _ in range(3):
            _lib.check(ctx.lib.pk_matpow(ctx.handle, op.handle, k, _ptr(U), _ptr(V)), "pk_matpow")
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(5):
            _lib.check(ctx.lib.pk_matpow(ctx.handle, op.handle, k, _ptr(U), _ptr(V)), "pk_matpow")
        ctx.sync()
        out["matpow_ms"] = (time.perf_counter() - t0) * 1e3 / 5
        nnz = int(val.numel())
        out["matpow_algorithmic_gb"] = ((8.0 * nnz if out["kernel"]["kernel"] == "dense-band" else 12.0 * nnz + 4.0 * (n + 1)) + (2 + 2 * k) * 8.0 * n) / 1e9
        out["matpow_gbs"] = out["matpow_algorithmic_gb"] / (out["matpow_ms"] * 1e-3)
        del U, V
    x, info = solve("kskipmrr", op, b, tol=1e-8, maxiter=400, use_graph=True, ctx=ctx, k=k)      # warm-up
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    its = 0
    for _ in range(2):
        x, info = solve("kskipmrr", op, b, tol=1e-8, maxiter=1200, use_graph=True, ctx=ctx, k=k)
        its += info["iterations"]
    e1.record()
    torch.cuda.synchronize()
    out["kskipmrr_iterations_per_s"] = its / (e0.elapsed_time(e1) * 1e-3)
    out["iterations"] = its
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""A short, un-graphed solve for `ncu` captures: one workload of bench.py, a few trips, plain stream launches.
    python tools/ncu_target.py kskipmrr8_p3d256 [maxiter]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PK_QUIET", "1")
import numpy as np
import torch
import bench
from parallel_krylov_b200 import device_problems as dp
from parallel_krylov_b200._core import Context, Operator, solve

name = sys.argv[1] if len(sys.argv) > 1 else "kskipmrr8_p3d256"
solver, k, kind, dims, cap = bench.WORKLOADS[name]
maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 3 * ((k or 0) + 1)
ctx = Context.get(0)
if kind == "stencil":
    rp, ci, va, n = dp.stencil_csr(*dims, ctx=ctx)
else:
    rp, ci, va, n = dp.banded_csr(dims[0], dims[1], 0, ctx=ctx)
op = Operator.from_csr_tensors(rp, ci, va, n, ctx)
b = dp.hash_normal(0, n, ctx=ctx)
kw = {"k": k} if k is not None else {}
x, info = solve(solver, op, b, tol=1e-8, maxiter=maxiter, use_graph=False, ctx=ctx, **kw)
torch.cuda.synchronize()
print(name, "iterations", info["iterations"], "launches", info["gpu_launches"], "residual", float(info["residual"][-1]))

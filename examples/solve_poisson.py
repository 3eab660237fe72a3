"""Single-GPU example: the five v3 entry points on a 3-D Poisson system (run on a B200).

    python examples/solve_poisson.py [n]         # grid n^3, default 128
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import parallel_krylov_b200 as pk
from parallel_krylov_b200 import device_problems as dp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
rowptr, col, val, N = dp.stencil_csr(n, n, n)             # CSR built directly in HBM
A = pk.Operator.from_csr_tensors(rowptr, col, val, N)     # or pass a scipy CSR / ndarray / (rowptr, col, val, N)
b = dp.hash_normal(0, N)

for name, kw in (("cg", {}), ("mrr", {}), ("kskipcg", {"k": 4}), ("kskipmrr", {"k": 8}), ("adaptivekskipmrr", {"k": 8})):
    x, info = getattr(pk, name)(A, b, tol=1e-8, maxiter=5000, **kw)       # prints the reference's banner
    true_res = float(torch.linalg.norm(b - A.matvec(x)) / torch.linalg.norm(b))
    print(f"  -> {info['iterations']} iterations, {info['iterations'] / info['time']:.0f} it/s, "
          f"true ||b-Ax||/||b|| = {true_res:.3e}\n")

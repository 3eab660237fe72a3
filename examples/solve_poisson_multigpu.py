"""Multi-GPU example (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 \
        examples/solve_poisson_multigpu.py 512

Each rank owns a contiguous block of rows (a z-slab of the grid) and the matching slice of every vector; a mat-vec
exchanges only the halo planes, dot products are all-reduced inside the kernels over NVLink.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from parallel_krylov_b200 import device_problems as dp
from parallel_krylov_b200 import mpi as pkm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
os.environ.setdefault("PK_QUIET", "0" if rank == 0 else "1")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = n ** 3
base = N // world
row0 = rank * base
n_loc = base if rank < world - 1 else N - row0
rowptr, col_global, val, _ = dp.stencil_csr(n, n, n, row0=row0, n_rows=n_loc)      # this rank's rows, global columns
local_A = pkm.DistOperator.from_local_csr(rowptr, col_global, val, N)             # halo plan + NCCL communicator
b_loc = dp.hash_normal(0, n_loc, offset=row0)

for name, kw in (("cg", {}), ("kskipmrr", {"k": 8})):
    x_loc, info = getattr(pkm, name)(None, local_A, b_loc, tol=1e-8, maxiter=5000, gather_x=False, **kw)
    if rank == 0:
        print(f"  -> {info['iterations']} iterations, {info['iterations'] / info['time']:.0f} it/s on {world} GPUs\n")
dist.destroy_process_group()

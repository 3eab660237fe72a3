"""Golden-vector self-check of the row-partitioned entry points, at whatever GPU count the process group has.

Every rank solves its contiguous row block of two fixed systems (3-D Poisson 48^3 and the banded-27 system with
N = 100 000) with ``cg``, ``mrr``, ``kskipcg`` k=2, ``kskipmrr`` k=4 and ``adaptivekskipmrr`` k=4 through
``parallel_krylov_b200.mpi.*`` — the path that replaces ``MultiGpu.dot`` of /root/reference/v3/gpu/mpi/common.py:138-165
and the loops of /root/reference/v3/gpu/mpi/cg.py:10-68 — and compares with the outputs of the UNMODIFIED reference
(``tests/golden/*.npz``, written by ``oracle/gen_golden.py`` from /root/reference/v3/cpu):

  * iteration count: +-2 (cg, mrr), +- max(k+1, 5 %) (k-skip variants);
  * residual history over the first 50 solver iterations: 1e-10 relative (k <= 2), 1e-8 relative + 1e-11 absolute (k = 4);
  * final TRUE residual ||b - A x|| / ||b|| of the gathered x below the tolerance.

bench.py runs this before its timed region at every N and puts the outcome into its JSON line (``"parity"``), so that
the multi-GPU parity is visible where the driver measures, not only on boxes with >= 2 GPUs under pytest.
``python -m torch.distributed.run --nproc-per-node N -m parallel_krylov_b200.selfcheck`` runs it stand-alone
(``__graft_entry__.smoke()`` does, with N = 2, when the box has two GPUs).

Only fixtures are read here; nothing under oracle/ is imported.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

SYSTEMS = {"p3d48": ("poisson3d", (48,)), "band27_100k": ("banded_spd", (100000, 13, 0))}
SOLVERS = (("cg", None), ("mrr", None), ("kskipcg", 2), ("kskipmrr", 4), ("adaptivekskipmrr", 4))
TOL = 1e-8


def case_id(mname, solver, k):
    return f"{solver}" + (f"_k{k}" if k is not None else "") + f"__{mname}__randn"


def compare(solver, k, gold, res, nosl, khist, true_relres):
    """Returns (list of failure strings, max relative deviation over the compared history entries)."""
    fails = []
    it, it_ref = int(nosl[-1]), int(gold["nosl"][-1])
    slack = 2 if k is None else max(k + 1, int(np.ceil(0.05 * it_ref)))
    if abs(it - it_ref) > slack:
        fails.append(f"iterations {it} vs reference {it_ref} (slack {slack})")
    m = min(len(res), len(gold["residual"]), int(np.searchsorted(gold["nosl"], 50, side="right")))
    ref = gold["residual"][:m]
    dev = float(np.max(np.abs(res[:m] - ref) / np.abs(ref))) if m else 0.0
    rtol, atol = (1e-10, 0.0) if (k or 0) <= 2 else (1e-8, 1e-11)
    if not np.all(np.abs(res[:m] - ref) <= atol + rtol * np.abs(ref)):
        fails.append(f"residual history deviates over the first {m} entries: max rel {dev:.3e} (rtol {rtol:g}, atol {atol:g})")
    if not np.array_equal(nosl[:m], gold["nosl"][:m]):
        fails.append("nosl differs")
    if not (true_relres < TOL * (1 + 1e-6)):
        fails.append(f"true residual {true_relres:.3e} >= tol")
    if khist is not None and "khistory" in gold:
        mk = min(len(khist), len(gold["khistory"]), m)
        if not np.array_equal(khist[:mk], gold["khistory"][:mk]):
            fails.append("khistory differs")
    return fails, dev


def run(group=None, verbose=False):
    """Collective over ``group`` (None = WORLD; torch.distributed must be initialised, world size >= 1).
    Returns the dict bench.py prints under "parity"."""
    import torch
    import torch.distributed as dist
    from . import mpi as pkm
    from . import problems

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    fails, dev_tight, dev_k4, n_cases = [], 0.0, 0.0, 0
    for mname, (kind, args) in SYSTEMS.items():
        A = problems.to_scipy(*getattr(problems, kind)(*args))
        n = A.shape[0]
        b = problems.rhs(n, "randn", 0)
        base = n // world
        lo = rank * base
        hi = n if rank == world - 1 else lo + base
        local = A[lo:hi]
        bnorm = float(np.linalg.norm(b))
        for solver, k in SOLVERS:
            cid = case_id(mname, solver, k)
            with np.load(os.path.join(GOLDEN, f"golden_{cid}.npz")) as z:
                gold = {key: z[key] for key in z.files}
            kw = {"k": k} if k is not None else {}
            x, info = getattr(pkm, solver)(group, local, b, tol=TOL, **kw)
            xh = x.cpu().numpy()
            tr = float(np.linalg.norm(b - A.dot(xh))) / bnorm
            kh = info["khistory"].cpu().numpy() if "khistory" in info else None
            f, dev = compare(solver, k, gold, info["residual"].cpu().numpy(), info["nosl"].cpu().numpy(), kh, tr)
            n_cases += 1
            if (k or 0) <= 2:
                dev_tight = max(dev_tight, dev)
            else:
                dev_k4 = max(dev_k4, dev)
            fails += [f"[{cid} rank {rank}/{world}] {msg}" for msg in f]
            if verbose and rank == 0:
                print(f"[selfcheck] {cid}: it={int(info['nosl'][-1])} (ref {int(gold['nosl'][-1])}) "
                      f"max_dev50={dev:.2e} true_res={tr:.3e} {'FAIL' if f else 'ok'}", file=sys.stderr, flush=True)
    dev_t = torch.tensor([len(fails), ], dtype=torch.float64)
    mx = torch.tensor([dev_tight, dev_k4], dtype=torch.float64)
    if world > 1:
        cdev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        dev_t, mx = dev_t.to(cdev), mx.to(cdev)
        dist.all_reduce(dev_t, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    n_fail = int(dev_t.item())
    for f in fails:
        print("PARITY FAIL", f, file=sys.stderr, flush=True)
    return {"cases": n_cases, "ranks": world, "max_dev50": float(mx[0].item()), "max_dev50_k4": float(mx[1].item()),
            "tolerance": {"k<=2": "1e-10 rel", "k=4": "1e-8 rel + 1e-11 abs", "iterations": "+-2 | +-max(k+1, 5%)",
                          "true_residual": "< 1e-8"},
            "systems": sorted(SYSTEMS), "solvers": [s + ("" if k is None else f" k={k}") for s, k in SOLVERS],
            "golden": "tests/golden (unmodified reference v3/cpu outputs)", "failures": n_fail, "ok": n_fail == 0,
            "first_failures": fails[:4]}


def main():
    import torch
    import torch.distributed as dist
    os.environ.setdefault("PK_QUIET", "1")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        init_single_rank_group()
    out = run(None, verbose=True)
    if rank == 0:
        print("SELFCHECK", "OK" if out["ok"] else "FAILED", json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if out["ok"] else 1


def init_single_rank_group():
    """A one-rank process group so that the mpi entry points can be exercised on a single GPU too."""
    import tempfile
    import torch.distributed as dist
    if dist.is_initialized():
        return
    fd, path = tempfile.mkstemp(prefix="pk_selfcheck_store_")
    os.close(fd)
    os.unlink(path)
    dist.init_process_group("gloo", store=dist.FileStore(path, 1), rank=0, world_size=1)


if __name__ == "__main__":
    sys.exit(main())

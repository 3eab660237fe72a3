"""Multi-GPU parity worker: launched by tests/test_gpu_dist.py (or by hand) as
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dist_worker.py
Every rank solves its row block through parallel_krylov_b200.mpi.* and checks the result against the oracle run on
the global system (same tolerances as the single-GPU parity tests)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
os.environ.setdefault("PK_QUIET", "1")

import numpy as np
import torch
import torch.distributed as dist

import krylov_oracle as oracle
from parallel_krylov_b200 import problems
from parallel_krylov_b200 import mpi as pkm


def _random_spd(n, seed):
    """Sparse symmetric, strictly diagonally dominant (=> SPD), columns spread over the whole index range."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    B = sp.random(n, n, density=6.0 / n, random_state=rng, format="csr", data_rvs=lambda m: -rng.uniform(0.1, 1.0, m))
    S = (B + B.T).tocsr()
    S.setdiag(0.0)
    S.eliminate_zeros()
    d = np.asarray(abs(S).sum(axis=1)).ravel() + 1.0
    A = (S + sp.diags(d)).tocsr()
    A.sort_indices()
    return A


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    failures = []
    mats = {
        "p3d20": problems.to_scipy(*problems.poisson3d(20)),
        "p3d_9x7x11": problems.to_scipy(*problems.poisson3d(9, 7, 11)),       # rows not divisible by the ranks
        "band27": problems.to_scipy(*problems.banded_spd(30011, 13, 0)),
        "dense": problems.dense_spd(384, 0),
        # unstructured: every rank exchanges with every other rank, send lists are not contiguous runs
        "random": _random_spd(6007, 0),
    }
    cases = [("cg", None), ("mrr", None), ("kskipcg", 2), ("kskipmrr", 2), ("kskipmrr", 4), ("adaptivekskipmrr", 4),
             ("cgcg", None)]
    for mname, A in mats.items():
        n = A.shape[0]
        base = n // world
        lo = rank * base
        hi = n if rank == world - 1 else lo + base
        local = A[lo:hi]
        b = problems.rhs(n, "randn", 0)
        for solver, k in cases:
            kw = {"k": k} if k is not None else {}
            xo, io = oracle.SOLVERS[solver](A, b.copy(), tol=1e-8, **kw)
            x, info = getattr(pkm, solver)(None, local, b, tol=1e-8, **kw)
            x = x.cpu().numpy()
            res = info["residual"].cpu().numpy()
            nosl = info["nosl"].cpu().numpy()
            tag = f"[{mname} {solver} k={k} rank {rank}]"
            try:
                assert x.shape[0] == n, "x must be the full-length solution on every rank"
                it, it_ref = int(nosl[-1]), int(io["nosl"][-1])
                slack = 2 if k is None else max(k + 1, int(np.ceil(0.05 * it_ref)))
                assert abs(it - it_ref) <= slack, (it, it_ref)
                m = min(len(res), len(io["residual"]), int(np.searchsorted(io["nosl"], 50, side="right")))
                rtol, atol = (1e-10, 0.0) if (k or 0) <= 2 else (1e-8, 1e-11)
                np.testing.assert_allclose(res[:m], io["residual"][:m], rtol=rtol, atol=atol)
                tr = oracle.true_relres(A, b, x)
                assert tr < 1e-8 * (1 + 1e-6), tr
                assert info["converged"]
            except AssertionError as e:
                failures.append(f"{tag} {e!r}"[:400])
    # ---- Jacobi-preconditioned Chronopoulos-Gear CG (one all-reduce per iteration), diagonal extracted per rank
    A = mats["band27"]; n = A.shape[0]; base = n // world; lo = rank * base; hi = n if rank == world - 1 else lo + base
    b = problems.rhs(n, "randn", 0)
    xo, io = oracle.cgcg(A, b.copy(), tol=1e-8, M=A.diagonal().copy())
    x, info = pkm.cgcg(None, A[lo:hi], b, tol=1e-8, M="jacobi")
    res = info["residual"].cpu().numpy()
    try:
        assert abs(int(info["nosl"][-1]) - int(io["nosl"][-1])) <= 2
        m = min(len(res), len(io["residual"]))
        np.testing.assert_allclose(res[:m], io["residual"][:m], rtol=1e-10)
        assert oracle.true_relres(A, b, x.cpu().numpy()) < 1e-8 * (1 + 1e-6)
    except AssertionError as e:
        failures.append(f"[cgcg jacobi rank {rank}] {e!r}"[:400])

    # ---- row-partitioned matrix powers (ghost rows of A from the neighbours, ONE exchange per trip): all k levels of both
    # chains must equal k chained global mat-vecs bit for bit on every rank's rows
    import ctypes as C
    from parallel_krylov_b200._core import _ptr
    from parallel_krylov_b200._lib import check
    A = mats["band27"]; n = A.shape[0]; base = n // world; lo = rank * base; hi = n if rank == world - 1 else lo + base
    op = pkm.DistOperator.from_any_local(A[lo:hi], None)
    if os.environ.get("PK_MATPOW", "1") not in ("0", "") and os.environ.get("PK_BAND_DIST", "1") not in ("0", ""):
        # band27 holds its full band: the k-skip MrR solves above ran their trips on the extended single-GPU operator
        # (dense-band matrix powers + fused steps, one ghost-zone exchange per trip)
        if getattr(op, "band_ext_op", None) is None:
            failures.append(f"[band27 rank {rank}] row-partitioned dense-band operator was not set up")
    if os.environ.get("PK_MATPOW", "1") not in ("0", ""):
        k = 8
        ld = op.ld
        rng = np.random.default_rng(9)
        u, v = rng.standard_normal(n), rng.standard_normal(n)
        U = torch.zeros((k + 1) * ld, dtype=torch.float64, device="cuda")
        V = torch.zeros((k + 1) * ld, dtype=torch.float64, device="cuda")
        U[:hi - lo] = torch.from_numpy(u[lo:hi]).cuda()
        V[:hi - lo] = torch.from_numpy(v[lo:hi]).cuda()
        torch.cuda.synchronize()
        try:
            assert op.matpow_ghost_rows >= 7 * 13, op.matpow_ghost_rows
            check(op.ctx.lib.pk_matpow(op.ctx.handle, op.handle, k, _ptr(U), _ptr(V)), "pk_matpow")
            op.ctx.sync()
            for l in range(1, k + 1):
                u, v = A.dot(u), A.dot(v)
                assert np.array_equal(U[l * ld:l * ld + hi - lo].cpu().numpy(), u[lo:hi]), f"chain 0 level {l}"
                assert np.array_equal(V[l * ld:l * ld + hi - lo].cpu().numpy(), v[lo:hi]), f"chain 1 level {l}"
        except Exception as e:
            failures.append(f"[distributed matrix powers rank {rank}] {e!r}"[:400])
    del op

    # ---- opt-in Chebyshev basis, row-partitioned: k = 8 must follow plain MrR (Gershgorin bounds all-reduced over the ranks)
    A = mats["p3d20"]; n = A.shape[0]; base = n // world; lo = rank * base; hi = n if rank == world - 1 else lo + base
    b = problems.rhs(n, "randn", 0)
    xo, io = oracle.mrr(A, b.copy(), tol=1e-8)
    x, info = pkm.kskipmrr(None, A[lo:hi], b, tol=1e-8, k=8, basis="chebyshev")
    nosl = info["nosl"].cpu().numpy(); res = info["residual"].cpu().numpy()
    try:
        it_mrr = int(io["nosl"][-1])
        assert it_mrr <= int(nosl[-1]) <= it_mrr + 9, (int(nosl[-1]), it_mrr)
        sel = nosl[nosl <= min(50, it_mrr)]
        np.testing.assert_allclose(res[:len(sel)], io["residual"][sel], rtol=1e-8)
        assert oracle.true_relres(A, b, x.cpu().numpy()) < 1e-8 * (1 + 1e-6)
    except AssertionError as e:
        failures.append(f"[chebyshev kskipmrr k=8 rank {rank}] {e!r}"[:400])

    # ---- structurally nonsymmetric A (upper block triangular across the ranks): the last rank references no remote column
    # but its rows are needed by its predecessor -> a pure sender must still send / push, and must be throttled by its
    # receiver during the back-to-back basis SpMVs of the k-skip variants
    import scipy.sparse as sp
    n = 20011
    tri = sp.diags([np.full(n, 4.0), np.full(n - 1, -1.0), np.full(n - 7, -0.5), np.full(n - 40, -0.25)], [0, 1, 7, 40],
                   format="csr")
    base = n // world; lo = rank * base; hi = n if rank == world - 1 else lo + base
    b = problems.rhs(n, "randn", 0)
    # (the k-skip recurrences assume a symmetric A — the reference itself diverges on this system — so: MrR, plus chained
    # two-vector operator applications with NO reduction in between, the pattern of the k-skip basis)
    xo, io = oracle.mrr(tri, b.copy(), tol=1e-8)
    x, info = pkm.mrr(None, tri[lo:hi], b, tol=1e-8)
    res = info["residual"].cpu().numpy()
    try:
        assert abs(int(info["nosl"][-1]) - int(io["nosl"][-1])) <= 2, (int(info["nosl"][-1]), int(io["nosl"][-1]))
        m = min(len(res), len(io["residual"]))
        np.testing.assert_allclose(res[:m], io["residual"][:m], rtol=1e-10)
        assert oracle.true_relres(tri, b, x.cpu().numpy()) < 1e-8 * (1 + 1e-6)
    except AssertionError as e:
        failures.append(f"[upper-triangular mrr rank {rank}] {e!r}"[:400])
    op = pkm.DistOperator.from_any_local(tri[lo:hi], None)
    v0, v1 = b.copy(), np.cos(np.arange(n, dtype=np.float64))
    u0 = torch.from_numpy(v0[lo:hi]).cuda()
    u1 = torch.from_numpy(v1[lo:hi]).cuda()
    for _ in range(24):                              # 24 exchanges back to back, both vectors per exchange
        u0, u1 = op.matvec(u0, x1=u1)
        u0, u1 = u0 * 0.25, u1 * 0.25
        v0, v1 = tri.dot(v0) * 0.25, tri.dot(v1) * 0.25
    if not (np.array_equal(u0.cpu().numpy(), v0[lo:hi]) and np.array_equal(u1.cpu().numpy(), v1[lo:hi])):
        failures.append(f"[upper-triangular chained two-vector SpMV rank {rank}] differs from scipy (must be bit-identical)")
    del op

    # ---- uneven blocks of ~1.0 M and ~1.4 M rows (2 ranks): the host's poll batch must be the same on every rank, or the
    # host-enqueued collectives of the NCCL paths lose their partner and the solve hangs
    if world == 2:
        A = problems.to_scipy(*problems.poisson3d(134))
        n = A.shape[0]
        cut = 1_000_000
        lo, hi = (0, cut) if rank == 0 else (cut, n)
        b = problems.rhs(n, "randn", 0)
        xo, io = oracle.cg(A, b.copy(), tol=1e-8, maxiter=60)
        x, info = pkm.cg(None, A[lo:hi], b, tol=1e-8, maxiter=60)
        res = info["residual"].cpu().numpy()
        try:
            assert int(info["nosl"][-1]) == int(io["nosl"][-1]) == 60
            np.testing.assert_allclose(res[:50], io["residual"][:50], rtol=1e-10)
        except AssertionError as e:
            failures.append(f"[uneven 1.0M/1.4M rank {rank}] {e!r}"[:400])

    # a vector given as this rank's slice only, an initial guess, and gather_x=False
    A = mats["p3d20"]; n = A.shape[0]; base = n // world; lo = rank * base; hi = n if rank == world - 1 else lo + base
    b = problems.rhs(n, "randn", 0); x0 = np.random.default_rng(3).standard_normal(n)
    xo, io = oracle.cg(A, b, x0.copy(), tol=1e-8)
    xs, info = pkm.cg(None, A[lo:hi], b[lo:hi], x=x0[lo:hi], tol=1e-8, gather_x=False)
    if xs.shape[0] != hi - lo or not np.allclose(xs.cpu().numpy(), xo[lo:hi], rtol=1e-6, atol=1e-9):
        failures.append(f"[slice inputs rank {rank}] mismatch")
    if abs(int(info["nosl"][-1]) - int(io["nosl"][-1])) > 2:
        failures.append(f"[slice inputs rank {rank}] iterations {int(info['nosl'][-1])} vs {int(io['nosl'][-1])}")

    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    for f in failures:
        print("FAIL", f, flush=True)
    if rank == 0:
        print("DIST_PARITY", "OK" if flag.item() == 0 else f"FAILED ({int(flag.item())})", f"world={world}",
              f"halo={os.environ.get('PK_HALO', 'p2p')} allreduce={os.environ.get('PK_ALLREDUCE', 'p2p')}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()

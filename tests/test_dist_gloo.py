"""World-size-2/3 `gloo` tests (CPU) of the host logic of the multi-GPU path: block-row partition, halo plan
(who sends which entries to whom, column renumbering, interior run) — validated by emulating the exchange with
gloo point-to-point and checking the assembled local mat-vec against the global one, bit for bit."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from parallel_krylov_b200 import problems


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _global_matrix(kind):
    if kind == "p3d":
        return problems.to_scipy(*problems.poisson3d(6, 5, 8))
    if kind == "band":
        return problems.to_scipy(*problems.banded_spd(301, 13, 0))
    if kind == "random":
        rng = np.random.default_rng(0)
        m = sp.random(257, 257, density=0.03, random_state=rng, format="csr") + sp.eye(257, format="csr")
        m.sort_indices()
        return m.tocsr()
    if kind == "blockdiag":       # no halo at all
        return sp.block_diag([problems.to_scipy(*problems.poisson2d(6))] * 3, format="csr")
    raise ValueError(kind)


def _split(n, world, uneven):
    if not uneven:
        base = n // world
        counts = [base] * world
        counts[-1] += n - base * world
    else:
        counts = [n // world + (7 if r == 0 else 0) for r in range(world)]
        counts[-1] = n - sum(counts[:-1])
    offs = np.concatenate([[0], np.cumsum(counts)])
    return offs


def _worker(rank, world, port, kind, uneven, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from parallel_krylov_b200.mpi._dist import build_halo_plan, row_offsets_from_local
        A = _global_matrix(kind)
        n = A.shape[0]
        offs = _split(n, world, uneven)
        lo, hi = int(offs[rank]), int(offs[rank + 1])
        loc = A[lo:hi].tocsr()
        loc.sort_indices()
        assert row_offsets_from_local(hi - lo) == [int(o) for o in offs]
        rowptr = torch.from_numpy(loc.indptr.astype(np.int32))
        colg = torch.from_numpy(loc.indices.astype(np.int64))
        plan = build_halo_plan(rowptr, colg, [int(o) for o in offs], rank)
        n_rows, n_halo = hi - lo, plan["n_halo"]
        # emulate the exchange pk_comm_halo_start performs (ncclSend/Recv there, gloo isend/irecv here)
        x = np.random.default_rng(5).standard_normal(n)
        x_ext = np.zeros(n_rows + n_halo)
        x_ext[:n_rows] = x[lo:hi]
        send_idx = plan["send_idx"].numpy()
        reqs, bufs = [], []
        for p in range(world):
            if p == rank:
                continue
            s0, s1 = plan["send_off"][p], plan["send_off"][p + 1]
            if s1 > s0:
                t = torch.from_numpy(x_ext[send_idx[s0:s1]].copy())
                reqs.append(dist.isend(t, p))
                bufs.append(t)
            r0, r1 = plan["recv_off"][p], plan["recv_off"][p + 1]
            if r1 > r0:
                t = torch.zeros(r1 - r0, dtype=torch.float64)
                reqs.append(dist.irecv(t, p))
                bufs.append((t, r0, r1))
        for r in reqs:
            r.wait()
        for b in bufs:
            if isinstance(b, tuple):
                x_ext[n_rows + b[1]: n_rows + b[2]] = b[0].numpy()
        # halo tail must be exactly the referenced external entries, in global order
        assert np.array_equal(x_ext[n_rows:], x[plan["halo_global"].numpy()])
        loc2 = sp.csr_matrix((loc.data, plan["col_local"].numpy(), loc.indptr), shape=(n_rows, n_rows + n_halo))
        y = loc2.dot(x_ext)
        assert np.array_equal(y, A.dot(x)[lo:hi])
        # interior rows reference owned columns only
        i0, i1 = plan["interior"]
        cl = plan["col_local"].numpy()
        assert np.all(cl[loc.indptr[i0]:loc.indptr[i1]] < n_rows)
        if kind == "blockdiag" and not uneven:
            assert n_halo == 0 and (i0, i1) == (0, n_rows)
        if kind == "p3d" and world == 2 and not uneven:
            assert n_halo == 30          # one 6x5 plane from the single neighbour
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,world,uneven", [("p3d", 2, False), ("band", 2, False), ("random", 2, True),
                                               ("blockdiag", 3, False), ("p3d", 3, True)])
def test_halo_plan_gloo(kind, world, uneven):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, uneven, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in results), results

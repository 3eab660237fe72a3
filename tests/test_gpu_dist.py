"""Multi-GPU parity (NCCL halo exchange + all-reduced dots) — needs >= 2 GPUs on the box; skipped otherwise.
The N>1 host logic is covered on CPU by tests/test_dist_gloo.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


# every exchange path stays green: the default (halo push fused into the SpMV kernel + in-kernel mailbox all-reduce over
# NVLink), the NCCL halo with the mailbox all-reduce, and NCCL for both
PATHS = [{}, {"PK_HALO": "nccl"}, {"PK_HALO": "nccl", "PK_ALLREDUCE": "nccl"}]


@pytest.mark.parametrize("world,paths", [(2, PATHS[0]), (2, PATHS[1]), (2, PATHS[2]), (4, PATHS[0])],
                         ids=["2-default", "2-nccl-halo", "2-nccl-both", "4-default"])
def test_distributed_parity(world, paths):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py")]
    env = dict(os.environ)
    env.update(paths)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "DIST_PARITY OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_selfcheck_golden_vectors(world):
    """bench.py's parity probe (golden vectors of the unmodified reference) at this GPU count."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), "-m", "parallel_krylov_b200.selfcheck"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "SELFCHECK OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])

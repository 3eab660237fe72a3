"""parallel_krylov_b200 — B200-native (sm_100a CUDA + NCCL) implementation of the v3 solver entry points of
5enxia/parallel-krylov: ``cg``, ``mrr``, ``kskipcg``, ``kskipmrr``, ``adaptivekskipmrr`` (single GPU here,
``parallel_krylov_b200.mpi`` for the row-partitioned multi-GPU variants).

    from parallel_krylov_b200.cg import cg          # ≙ from v3.gpu.cg import cg
    x, info = cg(A, b, tol=1e-8, maxiter=2000)
"""
from .cg import cg
from .mrr import mrr
from .kskipcg import kskipcg
from .kskipmrr import kskipmrr
from .adaptivekskipmrr import adaptivekskipmrr
from .cgcg import cgcg
from ._core import Context, Operator, PkError

__all__ = ["cg", "mrr", "kskipcg", "kskipmrr", "adaptivekskipmrr", "cgcg", "Context", "Operator", "PkError"]
